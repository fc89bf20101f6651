-- |
-- Module      : Data.TextCompression.Symbols
-- Description : host-side rank compression of arbitrary alphabets onto the byte symbols the kernels take
--
-- The reference's 'Data.BWT.toBWT' / 'Data.BWT.fromBWT' are polymorphic in @Ord a@ and its MTF / RLE kernels in
-- any @Pack b@ item (src/Data/BWT.hs:55,93; src/Data/RLE/Internal.hs:66-90).  Only order (BWT, MTF alphabet) and
-- equality (RLE) of the elements are ever used, so an input with at most 256 distinct elements is mapped to
-- dense 'Word8' ranks in alphabet order -- an order isomorphism -- sent through the same B200 kernels as a
-- ByteString, and mapped back.  Inputs with more distinct elements stay on the host (the kernels are 8-bit).
--
-- NOTE: written without a Haskell toolchain in the image (see Data.TextCompression.B200).
module Data.TextCompression.Symbols
  ( Alphabet
  , alphabetOf
  , alphabetSize
  , encode
  , decode
  , encodeBS
  ) where

import qualified Data.ByteString as BS
import           Data.Foldable   (toList)
import qualified Data.Map.Strict as M
import           Data.Sequence   (Seq)
import qualified Data.Sequence   as DS
import qualified Data.Set        as S
import           Data.Word       (Word8)

-- | The sorted distinct elements of an input and their ranks.
data Alphabet a = Alphabet (M.Map a Word8) (Seq a)

-- | 'Nothing' when the input has more than 256 distinct elements.
alphabetOf :: (Foldable f, Ord a) => f a -> Maybe (Alphabet a)
alphabetOf xs
  | S.size distinct > 256 = Nothing
  | otherwise             = Just (Alphabet (M.fromDistinctAscList (zip sorted [0 ..])) (DS.fromList sorted))
  where
    distinct = S.fromList (toList xs)
    sorted   = S.toAscList distinct

alphabetSize :: Alphabet a -> Int
alphabetSize (Alphabet _ back) = DS.length back

encode :: Ord a => Alphabet a -> a -> Word8
encode (Alphabet m _) x = m M.! x

decode :: Alphabet a -> Word8 -> a
decode (Alphabet _ back) w = DS.index back (fromIntegral w)

-- | The whole input as the ByteString of its ranks.
encodeBS :: (Foldable f, Ord a) => Alphabet a -> f a -> BS.ByteString
encodeBS al = BS.pack . map (encode al) . toList
