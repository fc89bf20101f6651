{-# LANGUAGE DeriveGeneric #-}

-- |
-- Module      :  Data.MTF.Internal
-- Description :  drop-in replacement of text-compression's Data.MTF.Internal over the B200 kernels
--
-- Export list and types of the reference (src/Data/MTF/Internal.hs:43-51).  'seqToMTF' returns the indices
-- and the FINAL list (not the initial alphabet), 'seqFromMTF' starts from the sorted final list, exactly like
-- the reference (SURVEY.md 2.3 Q5).  Items that can be told apart by bytes ('symbolBytes') run on the device
-- (tc_mtf_encode / tc_mtf_decode: chunked replay with composable summaries); anything else on the host.
-- NOT COMPILED: no GHC exists in the build image (see "Data.TextCompression.B200").
module Data.MTF.Internal ( -- * Base MTF types
                           MTF(..),
                           -- * Auxiliary functions
                           nubSeq',
                           -- * To MTF functions
                           seqToMTF,
                           -- * From MTF functions
                           seqFromMTF,
                         ) where

import           Data.Foldable             (foldl', toList)
import           Data.Maybe                (fromJust)
import           Data.RLE.Internal         (Pack, symbolBytes)
import           Data.Sequence             (Seq (..), (<|), (|>))
import qualified Data.Sequence             as DS
import qualified Data.Set                  as DSet
import qualified Data.TextCompression.B200 as B200
import           GHC.Generics              (Generic)

newtype MTF b = MTF ((Seq Int,Seq (Maybe b)))
  deriving (Eq,Ord,Show,Read,Generic)

-- | Sorted distinct elements ('Nothing' first): the initial move-to-front list.
nubSeq' :: Ord a => Seq (Maybe a) -> Seq (Maybe a)
nubSeq' = DS.fromList . DSet.toAscList . DSet.fromList . toList

moveToFront :: Int -> Seq a -> Seq a
moveToFront i l = DS.index l i <| DS.deleteAt i l

seqToMTF :: (Pack b, Ord b) => Seq (Maybe b) -> (Seq Int,Seq (Maybe b))
seqToMTF DS.Empty = (DS.empty, DS.empty)
seqToMTF xs = case symbolBytes xs of
  Just (ws, back) -> let (is, fl) = B200.seqToMTFW8 ws in (is, fmap (fmap back) fl)
  Nothing         -> foldl' step (DS.empty, nubSeq' xs) xs
  where
    step (out, l) y = let i = fromJust (DS.elemIndexL y l) in (out |> i, moveToFront i l)

seqFromMTF :: (Pack b, Ord b) => MTF b -> Seq (Maybe b)
seqFromMTF (MTF (DS.Empty,_)) = DS.empty
seqFromMTF (MTF (_,DS.Empty)) = DS.empty
seqFromMTF (MTF (is,fl)) = case symbolBytes fl of
  Just (ws, back) -> fmap (fmap back) (B200.seqFromMTFW8 (is, ws))
  Nothing         -> fst (foldl' step (DS.empty, nubSeq' fl) is)
  where
    step (out, l) i = (out |> DS.index l i, moveToFront i l)
