{-# LANGUAGE DeriveGeneric #-}

-- |
-- Module      :  Data.BWT
-- Description :  drop-in replacement of text-compression's Data.BWT over the B200 kernels
--
-- Export list and types of the reference (src/Data/BWT.hs:26-37).  'bytestringToBWT' / 'textToBWT' and the
-- ByteString / Text inverses go straight to the device; the polymorphic 'toBWT' / 'fromBWT' reach the same
-- kernels through rank compression when the input has at most 256 distinct elements.
-- NOT COMPILED: no GHC exists in the build image (see "Data.TextCompression.B200").
module Data.BWT ( -- * To BWT functions
                  toBWT,
                  bytestringToBWT,
                  TextBWT(..),
                  textToBWT,
                  -- * From BWT functions
                  fromBWT,
                  bytestringFromWord8BWT,
                  bytestringFromByteStringBWT,
                  textFromBWT,
                  tests
                ) where

import           Data.BWT.Internal

import           Data.ByteString              (ByteString)
import qualified Data.ByteString              as BS
import           Data.Foldable                (toList)
import           Data.Maybe                   (catMaybes)
import qualified Data.Sequence                as DS
import           Data.Text                    (Text)
import qualified Data.Text.Encoding           as DTE
import qualified Data.TextCompression.B200    as B200
import qualified Data.TextCompression.Symbols as Sym
import           Data.Word                    (Word8)
import           GHC.Generics                 (Generic)
import           Test.HUnit

-- | Burrows-Wheeler transform of any list with an order.
toBWT :: Ord a => [a] -> BWT a
toBWT [] = BWT DS.Empty
toBWT xs = case Sym.alphabetOf xs of
  Just al -> BWT (fmap (fmap (Sym.decode al)) (B200.toBWTW8 (Sym.encodeBS al xs)))
  Nothing -> let s = DS.fromList xs in BWT (saToBWT (createSuffixArray s) s)

bytestringToBWT :: ByteString -> BWT Word8
bytestringToBWT = BWT . B200.toBWTW8

newtype TextBWT = TextBWT (BWT Word8)
  deriving (Eq,Ord,Show,Read,Generic)

-- | The BWT of the UTF-8 bytes of the text.
textToBWT :: Text -> TextBWT
textToBWT = TextBWT . bytestringToBWT . DTE.encodeUtf8

-- | Inverse transform.  Like the reference: a column without 'Nothing' gives @[]@, one whose walk runs
-- into a second 'Nothing' throws @Maybe.fromJust: Nothing@.
fromBWT :: Ord a => BWT a -> [a]
fromBWT (BWT col)
  | DS.null col = []
  | otherwise   = case Sym.alphabetOf (catMaybes (toList col)) of
      Just al -> map (Sym.decode al) (BS.unpack (B200.fromBWTW8 (fmap (fmap (Sym.encode al)) col)))
      Nothing -> toList (magicInverseBWT (DS.unstableSortBy sortTB (DS.zip col (DS.fromList [0 .. DS.length col - 1]))))

bytestringFromWord8BWT :: BWT Word8 -> ByteString
bytestringFromWord8BWT (BWT col) = B200.fromBWTW8 col

bytestringFromByteStringBWT :: BWT ByteString -> ByteString
bytestringFromByteStringBWT = BS.concat . fromBWT

textFromBWT :: TextBWT -> Text
textFromBWT (TextBWT x) = DTE.decodeUtf8 (bytestringFromWord8BWT x)

-- | The reference has no direct BWT test (src/Data/BWT.hs:127-128).
tests :: Test
tests = TestList []
